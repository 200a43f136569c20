// sec_model.cpp — analysis tool (not product, not test): BASELINE config 4 (primary + shadow + 4 AO rays) through the
// kernel's own per-ray code (trace_core.cuh, host build); per pixel and stage it records the lean_step trips, then
// models how many warp trips different schedules need. Quoted in profiles/README.md.
//   python tools/model/dump_scene.py 12 && g++ -O2 -std=c++17 -ffp-contract=off -DYV_TEST_HOST_BUILD -o tools/model/_data/sm tools/model/sec_model.cpp && tools/model/_data/sm
#include <cstdio>
#include <vector>
#include <algorithm>
#include "../../yoxel-voxel_b200/csrc/trace_core.cuh"
using namespace yv;
struct St { U4 a[kMaxStack], b[kMaxStack]; void push(int sp,const U4&x,const U4&y){a[sp]=x;b[sp]=y;} void pop(int sp,U4&x,U4&y)const{x=a[sp];y=b[sp];} };
struct Fetch { const Rec*recs;
  void node(uint32_t idx,bool,uint32_t&m,uint32_t&cb)const{m=recs[idx].masks;cb=recs[idx].child_base;}
  uint32_t child_index(uint32_t,uint32_t cb,uint32_t m,uint32_t c)const{return cb+(uint32_t)__builtin_popcount((m>>8)&((1u<<c)-1u));}
  uint32_t root_index()const{return 0;} };
static int run(LeanState&s,const Fetch&f,St&st,bool front,bool&hit){int t=0;for(;;){int r=lean_step<false>(s,f,st,front);t++;if(r==kStepContinue)continue;hit=(r==kStepHit);return t;}}
int main(){
  FILE*f=fopen("tools/model/_data/recs.bin","rb"); fseek(f,0,SEEK_END); long n=ftell(f)/16; fseek(f,0,SEEK_SET);
  std::vector<Rec> recs(n); if(fread(recs.data(),16,n,f)!=(size_t)n)return 1; fclose(f);
  f=fopen("tools/model/_data/leaves.bin","rb"); fseek(f,0,SEEK_END); long nl=ftell(f)/4; fseek(f,0,SEEK_SET);
  std::vector<uint32_t> leaves(nl); if(fread(leaves.data(),4,nl,f)!=(size_t)nl)return 1; fclose(f);
  float cam[9]; f=fopen("tools/model/_data/cam.bin","rb"); if(fread(cam,4,9,f)!=9)return 1; fclose(f);
  const int W=1920,H=1080,S=6; const float pos[3]={0.5f,0.5f,0.3f}, light[3]={0.6f,0.4f,1.2f}; const float vox=1.0f/4096, aomax=0.05f;
  Fetch fetch{recs.data()}; St st;
  std::vector<int> L((size_t)W*H*S,0); std::vector<unsigned char> OCT((size_t)W*H*S,0);
  for(int y=0;y<H;y++)for(int x=0;x<W;x++){
    int*l=&L[((size_t)y*W+x)*S]; uint32_t pixel=y*W+x;
    float dx,dy,dz; primary_dir(cam,cam+3,cam+6,x,y,dx,dy,dz); dx=adjust_dir1(dx);dy=adjust_dir1(dy);dz=adjust_dir1(dz);
    LeanState s; bool hit=false;
    if(!lean_begin(s,fetch,true,pos[0],pos[1],pos[2],dx,dy,dz)) continue;
    l[0]=run(s,fetch,st,false,hit); if(!hit) continue;
    const Rec&r=recs[s.idx]; uint32_t c=s.ch^s.flags; uint32_t data=leaves[r.leaf_base+__builtin_popcount(r.masks&0xffu&((1u<<c)-1u))];
    float ht=max3f(s.t1x,s.t1y,s.t1z), nx,ny,nz; unpack_normal(data,nx,ny,nz);
    float Px=pos[0]+dx*ht,Py=pos[1]+dy*ht,Pz=pos[2]+dz*ht; float Ox=Px+nx*vox,Oy=Py+ny*vox,Oz=Pz+nz*vox;
    { float vx=light[0]-Ox,vy=light[1]-Oy,vz=light[2]-Oz; float len=sqrtf((vx*vx+vy*vy)+vz*vz);
      if(len>0){ float rx=adjust_dir1(vx/len),ry=adjust_dir1(vy/len),rz=adjust_dir1(vz/len); LeanState s2; if(lean_begin(s2,fetch,true,Ox,Oy,Oz,rx,ry,rz)){s2.tlimit=len; bool h; OCT[((size_t)y*W+x)*S+1]=(unsigned char)s2.flags; l[1]=run(s2,fetch,st,true,h);} } }
    for(int k=0;k<4;k++){ float ax,ay,az; ao_direction(nx,ny,nz,pixel,k,1,ax,ay,az); ax=adjust_dir1(ax);ay=adjust_dir1(ay);az=adjust_dir1(az);
      LeanState s2; if(lean_begin(s2,fetch,true,Ox,Oy,Oz,ax,ay,az)){s2.tlimit=aomax; bool h; OCT[((size_t)y*W+x)*S+2+k]=(unsigned char)s2.flags; l[2+k]=run(s2,fetch,st,true,h);} }
  }
  auto r4=[](int v){return (v+3)/4*4;};
  double lane=0; for(size_t i=0;i<L.size();i++) lane+=L[i];
  // A: lock-step rounds per 8x4 warp tile
  double A=0, A0=0; for(int ty=0;ty<H;ty+=4)for(int tx=0;tx<W;tx+=8) for(int s=0;s<S;s++){int mx=0; for(int j=0;j<4;j++)for(int i=0;i<8;i++){int yy=ty+j,xx=tx+i; if(yy<H&&xx<W) mx=std::max(mx,L[((size_t)yy*W+xx)*S+s]);} A+=r4(mx); if(s==0)A0+=r4(mx);}
  // A': per-lane back-to-back (no rounds): warp time = max over lanes of the sum
  double Ap=0; for(int ty=0;ty<H;ty+=4)for(int tx=0;tx<W;tx+=8){int mx=0; for(int j=0;j<4;j++)for(int i=0;i<8;i++){int yy=ty+j,xx=tx+i; if(yy<H&&xx<W){int sm=0; for(int s=0;s<S;s++) sm+=L[((size_t)yy*W+xx)*S+s]; mx=std::max(mx,sm);}} Ap+=r4(mx);}
  // B: primary as A; secondary over hit pixels compacted in 16x8-tile order (row-major inside the tile), lock-step rounds
  std::vector<size_t> hits; for(int ty=0;ty<H;ty+=8)for(int tx=0;tx<W;tx+=16)for(int j=0;j<8;j++)for(int i=0;i<16;i++){int yy=ty+j,xx=tx+i; if(yy<H&&xx<W){size_t p=(size_t)yy*W+xx; bool any=false; for(int s=1;s<S;s++) any|=L[p*S+s]>0; if(any) hits.push_back(p);}}
  double B=A0; for(size_t i=0;i<hits.size();i+=32) for(int s=1;s<S;s++){int mx=0; for(size_t j=i;j<std::min(hits.size(),i+32);j++) mx=std::max(mx,L[hits[j]*S+s]); B+=r4(mx);}
  // C: as B but the 5 secondary rays of a pixel run back to back in its lane (no rounds)
  double C=A0; for(size_t i=0;i<hits.size();i+=32){int mx=0; for(size_t j=i;j<std::min(hits.size(),i+32);j++){int sm=0; for(int s=1;s<S;s++) sm+=L[hits[j]*S+s]; mx=std::max(mx,sm);} C+=r4(mx);}
  // D: all secondary rays as one list (pixel-major in tile order), 32 per warp
  std::vector<int> rays; for(size_t p:hits) for(int s=1;s<S;s++) if(L[p*S+s]>0) rays.push_back(L[p*S+s]);
  double D=A0; for(size_t i=0;i<rays.size();i+=32){int mx=0; for(size_t j=i;j<std::min(rays.size(),i+32);j++) mx=std::max(mx,rays[j]); D+=r4(mx);}
  // E: secondary rays sorted by length inside groups of 1024 (an upper bound on what any regrouping of near-by rays can reach)
  double E=A0; for(size_t g=0;g<rays.size();g+=1024){ std::vector<int> v(rays.begin()+g, rays.begin()+std::min(rays.size(),g+1024)); std::sort(v.begin(),v.end()); for(size_t i=0;i<v.size();i+=32){int mx=0; for(size_t j=i;j<std::min(v.size(),i+32);j++) mx=std::max(mx,v[j]); E+=r4(mx);} }
  // F: the two-launch form: secondary rays binned by direction octant (dirFlags), inside a bin in tile order of the
  //    origin pixel, 32 per warp, lock-step
  double F=A0; size_t nF=0; for(int o=0;o<8;o++){ std::vector<int> v; for(size_t p:hits) for(int s=1;s<S;s++) if(L[p*S+s]>0 && OCT[p*S+s]==o) v.push_back(L[p*S+s]);
    nF+=v.size(); for(size_t i=0;i<v.size();i+=32){int mx=0; for(size_t j=i;j<std::min(v.size(),i+32);j++) mx=std::max(mx,v[j]); F+=r4(mx);} }
  // G: schedule A with every secondary round cut off after K trips; the rays that are not done are traced again from
  //    scratch in a second launch, compacted in tile order, lock-step (cost: their whole length again)
  double bestG=1e30; int bestK=0; double bestRe=0;
  for(int K=8;K<=96;K+=4){ double G=A0; std::vector<int> re;
    for(int ty=0;ty<H;ty+=4)for(int tx=0;tx<W;tx+=8) for(int s=1;s<S;s++){int mx=0; for(int j=0;j<4;j++)for(int i=0;i<8;i++){int yy=ty+j,xx=tx+i; if(yy<H&&xx<W){int v=L[((size_t)yy*W+xx)*S+s]; if(v>K) re.push_back(v); mx=std::max(mx,std::min(v,K));}} G+=r4(mx);}
    for(size_t i=0;i<re.size();i+=32){int mx=0; for(size_t j=i;j<std::min(re.size(),i+32);j++) mx=std::max(mx,re[j]); G+=r4(mx);}
    if(G<bestG){bestG=G;bestK=K;bestRe=(double)re.size();} }
  // histogram of secondary-ray lengths
  { std::vector<int> v=rays; std::sort(v.begin(),v.end()); printf("secondary ray trips: p10 %d p50 %d p90 %d p99 %d max %d mean %.1f\n", v[v.size()/10], v[v.size()/2], v[v.size()*9/10], v[v.size()*99/100], v.back(), [&]{double t=0; for(int x:v)t+=x; return t/v.size();}()); }
  double ideal=lane/32;
  printf("lane trips/pixel %.1f (primary %.1f); hit pixels %.1f%%; secondary rays %zu\n",lane/(W*H),A0*0+0.0,100.0*hits.size()/(W*H),rays.size());
  printf("warp trips (relative to ideal = lane trips / 32 = %.0f):\n",ideal);
  printf("  A  lock-step rounds per 8x4 tile (shipped) ........ %.0f  eff %.3f\n",A,ideal/A);
  printf("  A' same tiles, each lane runs its rays back to back  %.0f  eff %.3f\n",Ap,ideal/Ap);
  printf("  B  hit pixels compacted, lock-step rounds .......... %.0f  eff %.3f\n",B,ideal/B);
  printf("  C  hit pixels compacted, back to back .............. %.0f  eff %.3f\n",C,ideal/C);
  printf("  D  secondary rays as a list, 32 per warp ........... %.0f  eff %.3f\n",D,ideal/D);
  printf("  E  D with rays sorted by length in groups of 1024 .. %.0f  eff %.3f\n",E,ideal/E);
  printf("  F  D binned by direction octant (two-launch form) .. %.0f  eff %.3f  (%zu rays)\n",F,ideal/F,nF);
  printf("  G  A with rounds cut at K=%d trips, %.1f%% of the rays traced again compacted: %.0f  eff %.3f\n",bestK,100.0*bestRe/rays.size(),bestG,ideal/bestG);
  printf("  primary part of all of them: %.0f\n",A0);
  return 0; }
